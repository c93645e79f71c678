"""GPU parity of bags_loss.l1_loss_u8 (loss + gradient against the 8-bit ground truth in one kernel) with the
reference's operator sequence: PILtoTorch's `/ 255.0` (utils/general_utils.py) then l1_loss (utils/loss_utils.py:18-19)
and autograd."""
import pytest
import torch

from oracle import post_oracle as PO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 1080, 1920), (3, 37, 53), (1, 5, 7), (3, 840, 1297)], ids=str)
def test_l1_u8_matches_reference_sequence(shape):
    import bags_loss as BL
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape))
    img = torch.rand(*shape, generator=g).to(dev)
    gt = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8).to(dev)
    # the reference converts on the HOST (PILtoTorch: `torch.from_numpy(np.array(img)) / 255.0`, a true division;
    # torch's CUDA division by a scalar multiplies by the reciprocal and differs in the last bit for some values)
    gt_f = (gt.cpu().float() / 255.0).to(dev)
    img[..., :2, :] = gt_f[..., :2, :]                                  # exact ties: gradient must be exactly 0 there
    a = img.clone().requires_grad_(True)
    loss = BL.l1_loss_u8(a, gt)
    (loss * 3.0).backward()
    b = img.clone().requires_grad_(True)
    want = PO.l1_loss(b, gt_f)
    (want * 3.0).backward()
    w64 = PO.l1_loss(img.double(), gt_f.double())
    assert abs(float(loss.detach()) - float(w64)) <= 1e-6
    assert abs(float(loss.detach()) - float(want.detach())) <= 2e-6
    assert float((a.grad - b.grad).abs().max()) <= 3e-7 * float(b.grad.abs().max())   # +-3/n or 0
    assert torch.equal(a.grad == 0, b.grad == 0) and torch.equal(a.grad > 0, b.grad > 0)
    assert float(a.grad[..., :2, :].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        BL.l1_loss_u8(img.cpu(), gt.cpu())
    with pytest.raises(RuntimeError):
        BL.l1_loss_u8(img, gt.float())
