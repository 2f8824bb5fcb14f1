"""Import-path shim: `import cpp_wrappers.cpp_subsampling.grid_subsampling` (datasets/dataloader.py:5-6 in the reference) resolves to
the B200 implementation in apr_b200.cpp_wrappers."""
from apr_b200.cpp_wrappers.cpp_subsampling.grid_subsampling import *  # noqa: F401,F403
