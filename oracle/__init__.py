"""
CPU oracle package -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under multimesh_b200/ imports this package.

  oracle.capi      ctypes bindings of oracle/mm_oracle.c (canonical C restatement)
  oracle.np_oracle independent, naive numpy restatement (small cases; cross-checks the C)
  oracle.build     build recipes (oracle .so, and oracle/_ref from the reference's own C files)
"""
