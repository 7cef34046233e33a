"""Logger shim with the surface the hot-path modules use (reference utils/logging.py:44-120)."""

import logging
from typing import Union

FORMAT = "%(levelname).1s%(asctime)s.%(msecs)d %(process)d %(filename)s:%(lineno)d] %(message)s"
DATEFMT = "%m%d %H:%M:%S"

DEBUG: int = logging.DEBUG
INFO: int = logging.INFO
WARNING: int = logging.WARNING
ERROR: int = logging.ERROR

WHITE = "\x1b[37;20m"
WHITE_BOLD = "\x1b[37;1m"
BLUE = "\x1b[34;20m"
BLUE_BOLD = "\x1b[34;1m"
RED_BOLD = "\x1b[31;1m"
RESET = "\x1b[0m"

Logger = logging.Logger

_configured = False


def config_logging(*, fmt: str = FORMAT, level: Union[int, str] = logging.WARNING, datefmt: str = DATEFMT) -> None:
    """Configures the package logger once (the reference reconfigures the ROOT logger on every call)."""
    global _configured
    log = logging.getLogger("foundpose_b200")
    if not _configured:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(fmt, datefmt))
        log.addHandler(handler)
        log.propagate = False
        _configured = True
    log.setLevel(level)


def get_logger(level: int = logging.WARNING) -> Logger:
    config_logging(level=level)
    return logging.getLogger("foundpose_b200")


def get_separator(length: int = 80) -> str:
    return length * "-"


def log_heading(logger: Logger, msg: str, style: str = WHITE) -> None:
    separator = get_separator()
    logger.info(style + separator + RESET)
    logger.info(style + msg + RESET)
    logger.info(style + separator + RESET)
