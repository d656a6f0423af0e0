"""koafusion hot path on B200: hand-written sm_100a CUDA (tcgen05/TMEM/TMA) behind the reference's
``koafusion.models`` interface. ``oaprogressionmmf_b200.koamodels`` mirrors ``koafusion.models``;
``oaprogressionmmf_b200._lib`` binds the C ABI of include/koa_b200.h."""

__version__ = "0.1.0"
