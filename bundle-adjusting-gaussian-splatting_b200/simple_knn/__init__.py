"""Drop-in ``simple_knn`` (reference: simple-knn/, imported by scene/gaussian_model.py:20 as
``from simple_knn._C import distCUDA2``)."""
