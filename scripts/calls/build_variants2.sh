# Round 2, second A/B (run HERE before gpurun): what call 1 suggested
set -e
python -m sph_b200.build --variant packed2 -DSPH_PACKED=1 -DSPH_PACKED_RELAX=1
python -m sph_b200.build --variant trim -DSPH_TRIM=1
python -m sph_b200.build --variant packed_trim -DSPH_TRIM=1 -DSPH_PACKED=1 -DSPH_PACKED_RELAX=1
python -m sph_b200.build --variant g16s4 -DSPH_GRID_MULT=16 -DSPH_GRID_MULT_SORT=4
python -m sph_b200.build --variant g16s2 -DSPH_GRID_MULT=16 -DSPH_GRID_MULT_SORT=2
python -m sph_b200.build --variant g12s3 -DSPH_GRID_MULT=12 -DSPH_GRID_MULT_SORT=3
python -m sph_b200.build --variant pdl2 -DSPH_PDL=2
python -m sph_b200.build --variant pdl2_s4 -DSPH_PDL=2 -DSPH_GRID_MULT_SORT=4
python -m sph_b200.build --variant all1 -DSPH_TRIM=1 -DSPH_PACKED=1 -DSPH_PACKED_RELAX=1 -DSPH_GRID_MULT=16 -DSPH_GRID_MULT_SORT=4
python -m sph_b200.build --variant all2 -DSPH_TRIM=1 -DSPH_PACKED=1 -DSPH_PACKED_RELAX=1 -DSPH_GRID_MULT=16 -DSPH_GRID_MULT_SORT=4 -DSPH_PDL=2
