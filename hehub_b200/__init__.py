"""hehub_b200 — B200-native RNS polynomial-arithmetic backend for the HEhub ciphertext-op hot path.

The product is the CUDA shared library ``libhehub_b200.so`` (sources in ``csrc/``, C ABI in
``include/hehub_b200.h``) plus the C++ host mirror of the ``hehub::`` API in ``cpp/``.  This Python
package only carries the ctypes harness used by the tests and the bench (``binding.py``) and the
multi-GPU sweep driver (``sweep.py``).  There is no CPU implementation in this package.
"""
__all__ = ["binding"]
