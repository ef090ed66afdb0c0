"""smplpp_b200 — B200-native (sm_100a) implementation of SMPLpp's data-parallel hot path.

Host-side Python mirror of the reference's smplpp::SMPL / IkTask / VPoserDecoder interface over the C-ABI
shared library built from smplpp_b200/csrc (see include/smplpp_b200.h).  There is no CPU fallback: every
compute entry point raises when the CUDA extension is missing.
"""
