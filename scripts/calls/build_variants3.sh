# Round 2, third A/B (run HERE before gpurun)
set -e
F="-DSPH_PACKED=1 -DSPH_PACKED_RELAX=1 -DSPH_TRIM=1"
python -m sph_b200.build --variant nopipe -DSPH_PIPE=0
python -m sph_b200.build --variant pt $F
python -m sph_b200.build --variant pt_nopipe $F -DSPH_PIPE=0
python -m sph_b200.build --variant pt_g $F -DSPH_GRID_ADVECT=16 -DSPH_GRID_DENSITY=16
python -m sph_b200.build --variant onex $F -DSPH_ONE_EXCHANGE=1
