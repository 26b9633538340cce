"""Drop-in `utils` module: the hot-path helpers of the reference's utils.py on libaoslo_b200."""
from detecting_cones_aoslo_b200.utils import *  # noqa: F401,F403
