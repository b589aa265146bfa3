# Round 2, second half, A/B 4: n2 = n1 + 1024-cell scan tiles + k_advect on 16 blocks per SM; candidate rows from the sort key
set -e
S="-DSPH_SORT_SRC=1 -DSPH_SCAN_FAST=1 -DSPH_ASYNC=2 -DSCAN_ITEMS=4 -DSPH_GRID_ADVECT=16"
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant n2 $S
python -m sph_b200.build --variant n2_kr $S -DSPH_KEYROWS=1
python -m sph_b200.build --variant n2_kr_a3 $S -DSPH_KEYROWS=1 -DSPH_ASYNC=3
python -m sph_b200.build --variant n2_gd24 $S -DSPH_GRID_DENSITY=24
python -m sph_b200.build --variant n2_gd8 $S -DSPH_GRID_DENSITY=8
python -m sph_b200.build --variant n2_sg5 $S -DSPH_GRID_MULT_SORT=5
