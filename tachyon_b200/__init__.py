"""tachyon_b200 — B200-native variable-base MSM behind Tachyon's MSM-GPU C API."""
from .msm import MSMGpu  # noqa: F401
