"""Public surface of the package: the two reference classes, re-implemented on sm_100a kernels."""
from . import _lib  # noqa: F401
from . import attn_model  # noqa: F401
from . import hgru_module  # noqa: F401
from . import hgru_pose  # noqa: F401
from . import initialization  # noqa: F401
from . import pose_evaluation  # noqa: F401
from . import sharding  # noqa: F401
from . import tf_monkeydetector  # noqa: F401
from .hgru_module import ContextualCircuit, auxilliary_variables  # noqa: F401
from .hgru_pose import model  # noqa: F401
from .attn_model import attn_model_struct  # noqa: F401
from . import tf_checkpoint  # noqa: F401
from . import pipeline  # noqa: F401
from .pipeline import FramesToJoints, StreamedForward  # noqa: F401

__all__ = ["ContextualCircuit", "auxilliary_variables", "model", "attn_model_struct", "initialization",
           "hgru_module", "hgru_pose", "attn_model", "pose_evaluation", "sharding", "tf_monkeydetector",
           "tf_checkpoint", "pipeline", "FramesToJoints", "StreamedForward", "_lib"]
