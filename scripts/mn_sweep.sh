#!/bin/bash
# Record of the bring-up sweep for the MN-major (operands-in-place) tcgen05 weight gradient (round 1, B200):
# descriptor parameters were environment knobs at the time; of
#   {LBO 4096 B, SBO 512 B, K step 1024 B, layout SWIZZLE_128B_BASE32B, TMA SWIZZLE_128B_ATOM_32B}   <- derived from cute's Layout_MN_SW128_32B_Atom
# and seven perturbations of it (SBO 1024, LBO/SBO swapped, K step 512, layout SWIZZLE_128B with TMA SWIZZLE_128B, ...),
# only the derived combination passed tests/test_gpu_tc.py::test_conv_tc_wgrad (13/13); all others failed 13/13.
# The knobs were then frozen into constants (wgrad_tc.cu).  FSV_WGRAD_MN=0 still selects the planar re-layout path.
for mn in 1 0; do echo "== FSV_WGRAD_MN=$mn"; FSV_WGRAD_MN=$mn timeout 120 python -m pytest tests/test_gpu_tc.py -k wgrad -m gpu -q --timeout 60 -p no:cacheprovider 2>&1 | tail -1; done
