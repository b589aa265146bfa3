# Round 2, second half, A/B 6 (on the shipped defaults): deferred slot store in k_relax only, rare pairs out of line again
set -e
rm -f sph_b200/variants/*.so
python -m sph_b200.build --variant defer2 -DSPH_DEFER=2
python -m sph_b200.build --variant rare3 -DSPH_RELAX_RARE=1 -DSPH_RELAX_TRIP=3
python -m sph_b200.build --variant defer2_rare3 -DSPH_DEFER=2 -DSPH_RELAX_RARE=1 -DSPH_RELAX_TRIP=3
python -m sph_b200.build --variant defer2_rare2 -DSPH_DEFER=2 -DSPH_RELAX_RARE=1
