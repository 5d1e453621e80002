"""bialign_b200 -- B200-native bi-alignment engine behind the reference's `bialignment` surface.

    from bialign_b200 import bialignment          # drop-in for the reference module
    from bialign_b200.batch import BatchAligner   # many pairs, multi-GPU sharding

The DP runs only on the GPU (bialign_b200/libbialign_b200.so, built by `python -m bialign_b200.build`)."""
__all__ = ["bialignment", "batch", "encoding"]
