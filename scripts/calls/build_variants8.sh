# Round 2, second half, A/B 5: candidate rows of a thread's next particle prefetched into L1
set -e
S="-DSPH_SORT_SRC=1 -DSPH_SCAN_FAST=1 -DSCAN_ITEMS=4 -DSPH_GRID_ADVECT=16"
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant n2 $S -DSPH_ASYNC=2
python -m sph_b200.build --variant p4 $S -DSPH_ASYNC=6 -DSPH_PREFETCH=4
python -m sph_b200.build --variant p1 $S -DSPH_ASYNC=3 -DSPH_PREFETCH=1
python -m sph_b200.build --variant p2 $S -DSPH_ASYNC=2 -DSPH_PREFETCH=2
python -m sph_b200.build --variant p7 $S -DSPH_ASYNC=7 -DSPH_PREFETCH=7
