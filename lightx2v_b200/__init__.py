"""lightx2v_b200 — the LightX2V DiT denoising hot path (Wan2.1 / HunyuanVideo blocks, Wan + Hunyuan VAE decode) on hand-written sm_100a
kernels behind a C ABI (include/b200_dit.h -> csrc/libb200dit.so).

    lib        ctypes binding of every C entry point (raises B200Error; there is no CPU / torch fallback)
    host.*     Python classes with the reference's operator / infer-class interfaces (see host/__init__.py)

The CPU restatement used by the tests lives outside this package and is never imported from here (tests/test_host_cpu.py enforces it)."""

__version__ = "0.1.0"
