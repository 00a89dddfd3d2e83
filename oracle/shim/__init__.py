"""Import-name shim (torchsparse, torch_geometric) so the reference's models/*.py run
verbatim on CPU.  Put this directory on sys.path; it delegates to oracle/sparse_ref.py."""
