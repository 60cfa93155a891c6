"""opus_pllm_b200 — B200-native (sm_100a) implementation of the OPUS-PLLM protein-conditioned generation hot path.

Python here only mirrors the reference's interface (multi_modality_v1 encode / project / generate) and marshals
torch-owned device buffers into the C ABI of ``libopus_b200.so`` (include/opus_b200.h). No CPU fallback exists.
"""
__version__ = "0.1.0"
