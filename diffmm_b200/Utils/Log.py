"""Logger with the reference's interface (reference Utils/Log.py:7-39): ``Log(log_name, file_name)``
writes to logs/<time>_<file_name>.log (one shared file handler) and stdout; ``.info(msg)``."""
import logging
import os
import sys
from datetime import datetime
from typing import Optional


class Log():
    _shared_file_handler = None

    def __init__(self, log_name: str, file_name: Optional[str] = None):
        self.logger = logging.getLogger(log_name)
        self.logger.setLevel(logging.INFO)
        fmt = logging.Formatter('%(asctime)s - %(message)s', datefmt='%m/%d %H:%M:%S')
        if Log._shared_file_handler is None:
            try:
                os.makedirs("logs", exist_ok=True)
                log_time = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
                Log._shared_file_handler = logging.FileHandler(f"logs/{log_time}_{file_name or 'shared'}.log")
                Log._shared_file_handler.setFormatter(fmt)
            except OSError:            # read-only working directory: stdout only
                Log._shared_file_handler = None
        if Log._shared_file_handler is not None:
            self.logger.addHandler(Log._shared_file_handler)
        console_handler = logging.StreamHandler(sys.stdout)
        console_handler.setFormatter(fmt)
        self.logger.addHandler(console_handler)

    def info(self, message: str):
        self.logger.info(message)
