"""Mirror of the reference's ``model`` package namespace (model/__init__.py:1-4 + model/diff3.py)
for the classes on the inference path."""
from ..modules import GCRN, DiffUNet, DiffUNet1, DiffWave, Nocon, aia_complex_trans_ri  # noqa: F401
