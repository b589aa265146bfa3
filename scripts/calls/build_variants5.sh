# Round 2, second half, A/B 2: per-particle inputs staged one particle ahead with cp.async (SPH_ASYNC bit mask: 1 advect,
# 2 relax, 4 density), deferred arrival-slot store (SPH_DEFER); all on top of the source-order sort
set -e
S="-DSPH_SORT_SRC=1 -DSPH_SCAN_FAST=1"
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant sortsrc $S
python -m sph_b200.build --variant a1 $S -DSPH_ASYNC=1
python -m sph_b200.build --variant a2 $S -DSPH_ASYNC=2
python -m sph_b200.build --variant a3 $S -DSPH_ASYNC=3
python -m sph_b200.build --variant a7 $S -DSPH_ASYNC=7
python -m sph_b200.build --variant d $S -DSPH_DEFER=1
python -m sph_b200.build --variant a3d $S -DSPH_ASYNC=3 -DSPH_DEFER=1
python -m sph_b200.build --variant a7d $S -DSPH_ASYNC=7 -DSPH_DEFER=1
python -m sph_b200.build --variant a7d_rare2 $S -DSPH_ASYNC=7 -DSPH_DEFER=1 -DSPH_RELAX_RARE=1
python -m sph_b200.build --variant a7d_rare3 $S -DSPH_ASYNC=7 -DSPH_DEFER=1 -DSPH_RELAX_RARE=1 -DSPH_RELAX_TRIP=3
