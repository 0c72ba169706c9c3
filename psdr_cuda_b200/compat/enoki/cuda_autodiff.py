"""enoki.cuda / enoki.cuda_autodiff array types of the stand-in (the same classes: every array can carry a tangent)"""
from . import Float32, Matrix4f, Vector3f  # noqa: F401
