#!/bin/bash
# CTAs per SM of the two passes of the per-tensor RTN route (cfg1): B200Q_TENSOR_P_CTAS x B200Q_TENSOR_F_CTAS
for p in 4 2 3; do for f in 4 2 3; do echo "P=$p F=$f"; B200Q_TENSOR_P_CTAS=$p B200Q_TENSOR_F_CTAS=$f python tools/prof_small.py 2>&1 | grep "cfg1 raw\|cfg1 bench"; done; done
