from multimesh_b200.components import plotter as _impl

globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
