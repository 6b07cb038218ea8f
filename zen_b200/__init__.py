"""zen_b200: B200-native (sm_100a) implementation of Zen's HPR hot path.

Python here is only the host-side mirror used by tests/bench; the product is
the C-ABI shared library declared in include/zen_b200.h.
"""
