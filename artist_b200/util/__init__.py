"""Utilities: device selection and the distributed-environment contract of ``artist/util/env.py``."""
