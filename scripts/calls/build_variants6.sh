# Round 2, second half, A/B 3: on top of n1 = source-order sort + fast scan + staged k_relax inputs
set -e
S="-DSPH_SORT_SRC=1 -DSPH_SCAN_FAST=1 -DSPH_ASYNC=2"
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant n1 $S
python -m sph_b200.build --variant n1_pv4 $S -DSPH_ADVECT_PV4=1
python -m sph_b200.build --variant n1_scan4 $S -DSCAN_ITEMS=4
python -m sph_b200.build --variant n1_scan16 $S -DSCAN_ITEMS=16
python -m sph_b200.build --variant n1_sg4 $S -DSPH_GRID_MULT_SORT=4
python -m sph_b200.build --variant n1_sg6 $S -DSPH_GRID_MULT_SORT=6
python -m sph_b200.build --variant n1_sg12 $S -DSPH_GRID_MULT_SORT=12
python -m sph_b200.build --variant n1_rb5 $S -DSPH_BLOCKS_RELAX=5
python -m sph_b200.build --variant n1_ab5 $S -DSPH_BLOCKS_ADVECT=5
python -m sph_b200.build --variant n1_gr4 $S -DSPH_GRID_RELAX=4
python -m sph_b200.build --variant n1_gr16 $S -DSPH_GRID_RELAX=16
python -m sph_b200.build --variant n1_ga4 $S -DSPH_GRID_ADVECT=4
python -m sph_b200.build --variant n1_ga16 $S -DSPH_GRID_ADVECT=16
