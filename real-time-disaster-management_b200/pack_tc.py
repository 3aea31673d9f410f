"""Operand images for the tensor-core (tcgen05) kernels.  Filled in as those kernels land."""
from __future__ import annotations


def derive_tc(sd, arch, precision):
    return {}
