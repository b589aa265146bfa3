# Round 2, second half: A/B list of the restructurings written after the per-instruction counts of the r2_final capture
# (run HERE before gpurun; scripts/calls/gpu_r2b_call1.sh times them)
set -e
S="-DSPH_SORT_SRC=1 -DSPH_SCAN_FAST=1"
python -m sph_b200.build --variant scanfast -DSPH_SCAN_FAST=1
python -m sph_b200.build --variant sortsrc $S
python -m sph_b200.build --variant rare2 -DSPH_RELAX_RARE=1
python -m sph_b200.build --variant rare3 -DSPH_RELAX_RARE=1 -DSPH_RELAX_TRIP=3
python -m sph_b200.build --variant rare4 -DSPH_RELAX_RARE=1 -DSPH_RELAX_TRIP=4
python -m sph_b200.build --variant pm -DSPH_PAIRMASK=1
python -m sph_b200.build --variant all2 $S -DSPH_RELAX_RARE=1 -DSPH_PAIRMASK=1
python -m sph_b200.build --variant all3 $S -DSPH_RELAX_RARE=1 -DSPH_RELAX_TRIP=3 -DSPH_PAIRMASK=1
python -m sph_b200.build --variant all2_s16 $S -DSPH_RELAX_RARE=1 -DSPH_PAIRMASK=1 -DSPH_GRID_MULT_SORT=16
python -m sph_b200.build --variant all2_nopm $S -DSPH_RELAX_RARE=1
