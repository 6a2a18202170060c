"""Wall-clock context manager with the reference's log format (mirror of
pyminiweather/utils/timing.py:15-36).  Device work is asynchronous, so the block synchronises the
fields' CUDA stream on exit when given one."""
import logging
import time

logger = logging.getLogger("pyminiweather.log")


class TimedCodeBlock:
    def __init__(self, label: str = "Elapsed time", sync=None):
        self.label, self.sync = label, sync
        self.elapsed_time = 0.0

    def __enter__(self):
        self._t0 = time.perf_counter()
        return self._t0

    def __exit__(self, exc_type, exc, tb):
        if self.sync is not None and exc_type is None:
            self.sync()
        self.elapsed_time = time.perf_counter() - self._t0
        logger.info(f"{self.label}: {self.elapsed_time} s")
