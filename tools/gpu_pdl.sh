#!/bin/bash
# A/B of the programmatic-dependent-launch trigger position on small and large grids (chained column matters)
python tools/ab_bench.py 1024 500 "nopdl(default):" "pdl_pos0:GGP_PDL=1" "pdl_pos1:GGP_PDL=1,GGP_PDL_POS=1" "pdl_pos2:GGP_PDL=1,GGP_PDL_POS=2" "pdl_pos3:GGP_PDL=1,GGP_PDL_POS=3"
python tools/ab_bench.py 512 500 "nopdl(default):" "pdl_pos2:GGP_PDL=1,GGP_PDL_POS=2" "pdl_pos3:GGP_PDL=1,GGP_PDL_POS=3"
python tools/ab_bench.py 2048 300 "pos0(default):" "pos1:GGP_PDL_POS=1" "pos2:GGP_PDL_POS=2" "pos3:GGP_PDL_POS=3"
