"""Run pytest as if /root/reference did not exist (the situation on the GPU box): os.path.isdir / exists answer False for it.

    python tools/pytest_without_reference.py tests -q -m "not gpu"

The reference-executing tests must skip, everything else -- the committed fixtures they produced -- must still pass."""
import os
import sys

if __name__ == "__main__":
    _isdir, _exists = os.path.isdir, os.path.exists

    def _hidden(p):
        return isinstance(p, (str, os.PathLike)) and os.fspath(p).startswith("/root/reference")

    os.path.isdir = lambda p: False if _hidden(p) else _isdir(p)
    os.path.exists = lambda p: False if _hidden(p) else _exists(p)
    import pytest
    sys.exit(pytest.main(sys.argv[1:] + ["-p", "no:cacheprovider"]))
