"""Drop-in `image_transformation` module (crop_image, mirror_padding_image)."""
from detecting_cones_aoslo_b200.image_transformation import *  # noqa: F401,F403
