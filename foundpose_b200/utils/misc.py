"""The few helpers of the reference's utils/misc.py that the hot path touches (:30-45, 559-603)."""

import time
from typing import Optional

import numpy as np
import torch

from foundpose_b200.utils import logging

logger: logging.Logger = logging.get_logger()


class Timer:
    """Wall-clock timer (reference utils/misc.py:30-45).

    Unlike the reference, `elapsed` synchronises the current CUDA device first when
    `cuda_sync=True`, so stage times are attributed to the stage that launched the work.
    """

    def __init__(self, enabled: bool = True, cuda_sync: bool = False) -> None:
        self.enabled = enabled
        self.cuda_sync = cuda_sync
        self.start_time = None

    def start(self):
        if self.enabled:
            if self.cuda_sync and torch.cuda.is_available():
                torch.cuda.synchronize()
            self.start_time = time.time()

    def elapsed(self, msg="Elapsed") -> Optional[float]:
        if self.enabled:
            if self.cuda_sync and torch.cuda.is_available():
                torch.cuda.synchronize()
            elapsed = time.time() - self.start_time
            logger.info(f"{msg}: {elapsed:.5f}s")
            return elapsed
        else:
            return None


def array_to_tensor(array: np.ndarray, make_array_writeable: bool = True) -> torch.Tensor:
    if not array.flags.writeable:
        if make_array_writeable and array.flags.owndata:
            array.setflags(write=True)
        else:
            array = np.array(array)
    return torch.from_numpy(array)


def tensor_to_array(tensor: torch.Tensor) -> np.ndarray:
    return tensor.detach().cpu().numpy()
