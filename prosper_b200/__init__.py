"""prosper_b200 -- B200-native truncated-EM (Expectation Truncation) engine behind prosper's API.

Mirrors the reference's operator interface for ONE hot path (select_Hprimes -> E_step ->
M_step of prosper/em/camodels/*_et.py); see DESIGN.md.  The arithmetic runs in hand-written
sm_100a CUDA kernels reached through the C ABI of include/prosper_b200.h; there is no CPU
fallback -- constructing a model without the built library or without a B200 raises.
"""
__version__ = "0.1.0"
