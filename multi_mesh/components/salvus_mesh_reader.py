from multimesh_b200.components import salvus_mesh_reader as _impl

globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
