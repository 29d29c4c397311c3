from multimesh_b200.api import *  # noqa: F401,F403
from multimesh_b200 import api as _impl

globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
