from .base import *  # noqa: F401,F403
from .factory import get_calibrator  # noqa: F401
from .minmax import MinMaxCalibrator  # noqa: F401
