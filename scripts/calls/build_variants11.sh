# Round 2, second half, A/B 9: blocks per SM of the scan on its own (the default build before this call ran it on the 4 of the sort grid)
set -e
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant sc4 -DSPH_GRID_MULT_SCAN=4
python -m sph_b200.build --variant sc12 -DSPH_GRID_MULT_SCAN=12
python -m sph_b200.build --variant sc16 -DSPH_GRID_MULT_SCAN=16
python -m sph_b200.build --variant sc16_t2 -DSPH_GRID_MULT_SCAN=16 -DSCAN_ITEMS=2
