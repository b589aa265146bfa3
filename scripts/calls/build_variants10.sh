# Round 2, second half, A/B 8 (on the shipped defaults + straight-line relax walk): entries per trip of the sort kernels
set -e
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant si1 -DSPH_SORT_ITEMS=1
python -m sph_b200.build --variant si4 -DSPH_SORT_ITEMS=4
python -m sph_b200.build --variant si4_sg4 -DSPH_SORT_ITEMS=4 -DSPH_GRID_MULT_SORT=4
python -m sph_b200.build --variant si4_sg6 -DSPH_SORT_ITEMS=4 -DSPH_GRID_MULT_SORT=6
python -m sph_b200.build --variant si8_sg4 -DSPH_SORT_ITEMS=8 -DSPH_GRID_MULT_SORT=4
