"""Logging configuration (reference: utils/LoggingConfig.java + logging.properties with REEF's
ThreadLogFormatter; JobLogger prefixes live in jobserver/driver.py)."""
from __future__ import annotations

import logging
import os


def configure(level: str = None, test: bool = False) -> None:
    lvl = getattr(logging, (level or os.environ.get("HB_LOG_LEVEL", "WARNING" if test else "INFO")).upper())
    logging.basicConfig(level=lvl, format="%(asctime)s %(levelname)s %(threadName)s %(name)s | %(message)s",
                        force=True)
