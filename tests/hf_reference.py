"""Shim: the reference-arithmetic harness lives in oracle/hf_reference.py (test infrastructure, like the rest of oracle/)."""
from oracle.hf_reference import *  # noqa: F401,F403
from oracle.hf_reference import _construct_on, _cast_keep_float_buffers  # noqa: F401
