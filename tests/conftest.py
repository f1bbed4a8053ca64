import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The compiled reference (oracle/_ref/libref_mtrack.so) keeps O(27 * 8 * K) bytes of VLAs on the caller's stack
    # (global_tracker.cpp:319-324,609-611): ~7 MB at the 30 k keylines of the 1280x960 configuration,
    # which does not fit the default 8 MB.  Raise the soft limit of this process (the main thread's stack grows on demand).
    try:
        import resource
        soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
        want = 1 << 30
        if hard != resource.RLIM_INFINITY:
            want = min(want, hard)
        if soft == resource.RLIM_INFINITY or soft < want:
            resource.setrlimit(resource.RLIMIT_STACK, (want, hard))
    except (ImportError, ValueError, OSError):
        pass


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the CUDA library and the reference oracle."""
    import __graft_entry__ as g
    g.build()
    return True
