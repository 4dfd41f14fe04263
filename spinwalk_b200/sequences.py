"""Sequence tables for the bench / test workloads, restating the reference's text generators on plain values.

`pgse()` follows dMRI::pgse::run (src/dwi/pgse.cpp:67-153): two rectangular gradient lobes of duration delta whose
starts are DELTA apart, a 90 at t = 0 and a 180 (phase 90) midway, the gradient amplitude of the FIRST b-value and one
gradient scale sqrt(b_i / b_0) per b-value (WHAT_TO_SCALE = 1).  The reference writes these numbers into an ini with
std::to_string (6 decimals) and reads them back as float; the same rounding is applied here.
"""
from __future__ import annotations

import math

GAMMA = 267515315.0  # rad/s/T (src/definitions.h:20)


def _to_string(x: float) -> float:
    """std::to_string(double) -> "%f" (6 decimals) -> std::stof / istream >> float."""
    return float(f"{x:f}")


def pgse(b_values, direction=(1.0, 0.0, 0.0), start_ms=15, delta_ms=10, DELTA_ms=20, timestep_us=50):
    """Returns SimConfig keyword arguments: RF_*, gradient_* (microseconds / mT/m), scales, scale_type."""
    if DELTA_ms < delta_ms:
        raise ValueError("DELTA must be greater than delta")
    d, D = delta_ms * 1e-3, DELTA_ms * 1e-3
    G = math.sqrt(b_values[0] * 1e6 / (GAMMA * GAMMA * d * d * (D - d / 3.0))) * 1000.0  # mT/m (pgse.cpp:82-84)
    norm = math.sqrt(sum(v * v for v in direction))
    if norm == 0:
        raise ValueError("Direction vector is zero")
    u = [v / norm for v in direction]
    start_us = start_ms * 1000
    n = delta_ms * 1000 // timestep_us
    lobe = lambda g: [0.0] + [_to_string(G * g)] * n + [0.0]
    gx, gy, gz = (lobe(g) + lobe(g) for g in u)
    gap = (DELTA_ms - delta_ms) * 1000
    t = [start_us - timestep_us] + [start_us + i * timestep_us for i in range(n)] + [start_us + n * timestep_us + timestep_us]
    t += [start_us + gap + n * timestep_us - timestep_us] + [start_us + gap + i * timestep_us for i in range(n, 2 * n)] \
        + [start_us + gap + 2 * n * timestep_us + timestep_us]
    return dict(RF_FA_deg=[90.0, 180.0], RF_PH_deg=[0.0, 90.0],
                RF_T_us=[0, int(start_us + delta_ms * 1000 + (DELTA_ms - delta_ms) * 1000 / 2)],
                gradient_X_mTm=gx, gradient_Y_mTm=gy, gradient_Z_mTm=gz, gradient_T_us=t,
                scales=[_to_string(math.sqrt(b / b_values[0])) for b in b_values], scale_type=1)
