"""slam_llm_b200 — B200 (sm_100a) kernels + host engine for the SLAM-LLM training-step hot path.

Layout: csrc/ (CUDA kernels + C ABI, built into libslam_b200.so by build.py), lib.py (ctypes binding),
ops.py (tensor-level wrappers), engine.py (encoder / projector / decoder step on those kernels).
"""
__all__ = ["lib", "ops"]
