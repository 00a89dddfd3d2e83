"""ORACLE — test infrastructure only.  CPU restatement of the reference hot path
(InstanceRefer.forward and the third-party operators under it).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it."""
