"""B200-native drop-in for the per-image transform chain of Tezahc/image_processor_pipeline.

* ``ops``        in-memory operators on CUDA uint8 tensors (hand-written sm_100a kernels
                 behind the C ABI of ``include/ippgpu.h``);
* ``transforms`` the reference's process functions, same names / signatures / returns /
                 exceptions / file naming, with the arithmetic on the GPU;
* ``pipeline``   ``ProcessingStep`` / ``ProcessingPipeline`` with the reference's contract;
* ``chain``      the fused, batched, device-resident full chain + multi-GPU sharding.

There is no CPU fallback: importing works anywhere, computing needs libippgpu.so and a
CUDA device and raises otherwise.
"""
__version__ = "0.1.0"


def install_reference_alias():
    """Make this package importable as ``image_processor_pipeline`` (the reference's own
    package name, used by its modules' absolute imports — SURVEY.md §8b)."""
    import sys
    sys.modules.setdefault("image_processor_pipeline", sys.modules[__name__])
