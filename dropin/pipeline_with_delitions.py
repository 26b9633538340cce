"""`sys.path.insert(0, '<repo>/dropin')` makes `from pipeline_with_delitions import find_blobs`
resolve to the B200 implementation (same signature as the reference's module)."""
from detecting_cones_aoslo_b200.pipeline_with_delitions import *  # noqa: F401,F403
from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs  # noqa: F401
