# Round 2, second half, A/B 10: last small knobs on the shipped defaults
set -e
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant t4 -DSPH_RELAX_TRIP=4
python -m sph_b200.build --variant sg3 -DSPH_GRID_MULT_SORT=3
python -m sph_b200.build --variant sg5 -DSPH_GRID_MULT_SORT=5
python -m sph_b200.build --variant gr6 -DSPH_GRID_RELAX=6
python -m sph_b200.build --variant gr12 -DSPH_GRID_RELAX=12
