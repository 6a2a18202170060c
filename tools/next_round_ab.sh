#!/bin/bash
# Builds every experiment prepared at the end of round 1 ("Candidates for the next round")
# and prints the gpurun command that times them all against the default build in ONE call (about 10 s per
# variant on the box).  Build here (no GPU needed), run there.
set -e
cd "$(dirname "$0")/.."
python tools/build_variant.py def &
python tools/build_variant.py zhyb   -DPMW_ZSWEEP_UNIVERSAL=2 &
python tools/build_variant.py zuni   -DPMW_ZSWEEP_UNIVERSAL=1 &
python tools/build_variant.py xbgpf  -DPMW_XSWEEP_BGPF=1 &
wait
python tools/build_variant.py xrot3  -DPMW_XSWEEP_ROT3=1 &
python tools/build_variant.py x128   -DPMW_XSWEEP_ROT3=1 -DPMW_XSWEEP_MINB=4 &
python tools/build_variant.py x128l  -DPMW_XSWEEP_ROT3=1 -DPMW_XSWEEP_MINB=4 -DPMW_XSWEEP_LEAN=1 &
python tools/build_variant.py onepow -DPMW_STATS_ONEPOW=1 &
wait
echo
echo "gpurun --timeout 600 -- 'python tools/ab_sweeps.py 2048 1024 500 def zhyb zuni xbgpf xrot3 x128 x128l 2>&1 | tee gpurun_out/next_ab.log'"
echo "# a variant that wins: PMW_LIB=\$PWD/pyminiweather_b200/variants/libpmw_TAG.so python -m pytest tests -m gpu -x -q"
