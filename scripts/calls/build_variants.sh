# Build every tuning variant of scripts/calls/gpu_round2_first.sh into sph_b200/variants/ (run HERE, before gpurun: the .so
# files travel with the snapshot).  ~7 s each.
set -e
python -m sph_b200.build --variant packed -DSPH_PACKED=1
python -m sph_b200.build --variant packed_relax -DSPH_PACKED=1 -DSPH_PACKED_RELAX=1
python -m sph_b200.build --variant packed_b3 -DSPH_PACKED=1 -DSPH_BLOCKS_ADVECT=3
python -m sph_b200.build --variant pd4 -DSPH_RELAX_PD4=1
python -m sph_b200.build --variant relax_b3 -DSPH_BLOCKS_RELAX=3
python -m sph_b200.build --variant pdl -DSPH_PDL=1
python -m sph_b200.build --variant packed_pdl -DSPH_PACKED=1 -DSPH_PDL=1
python -m sph_b200.build --variant flat -DSPH_GRID_MULT=0
python -m sph_b200.build --variant m4 -DSPH_GRID_MULT=4
python -m sph_b200.build --variant m16 -DSPH_GRID_MULT=16
python -m sph_b200.build --variant packed_flat -DSPH_PACKED=1 -DSPH_GRID_MULT=0
python -m sph_b200.build --variant packed_pd4_pdl_flat -DSPH_PACKED=1 -DSPH_RELAX_PD4=1 -DSPH_PDL=1 -DSPH_GRID_MULT=0
