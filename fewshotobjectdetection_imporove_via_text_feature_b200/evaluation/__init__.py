from .calibration_layer import PrototypicalCalibrationBlock
from . import detection_formats  # noqa: F401
