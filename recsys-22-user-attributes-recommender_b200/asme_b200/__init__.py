"""asme_b200 -- B200-native (sm_100a) hot path for the ASME sequential recommenders.

Host side in Python (mirroring the reference's model / module / metric classes), arithmetic in
hand-written CUDA behind the C ABI of ``include/asme_b200.h`` (``lib/libasme_b200.so``).
Importing the package does not need a GPU; calling any op does, and fails loudly otherwise.
"""
__version__ = "0.1.0"
