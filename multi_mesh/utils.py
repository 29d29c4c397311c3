from multimesh_b200.utils import *  # noqa: F401,F403
from multimesh_b200 import utils as _impl

globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
