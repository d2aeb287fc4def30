"""Package logger ("onnx_quantize_b200"): INFO to stdout by default, adjustable with
``set_log_level`` — same entry point as the reference's ``_logging.py`` (:43-54)."""
import logging
import sys

__all__ = ["set_log_level"]

_NAME = "onnx_quantize_b200"


def _install_handler() -> None:
    log = logging.getLogger(_NAME)
    log.setLevel(logging.INFO)
    if not log.handlers:
        h = logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter("%(name)s - %(levelname)s - %(message)s"))
        log.addHandler(h)
        log.propagate = False


def set_log_level(level: int) -> None:
    """Set the verbosity of every logger of this package."""
    logging.getLogger(_NAME).setLevel(level)


_install_handler()
