#!/bin/bash
# ON THE GPU BOX: schedule timings of the bicycle batch under the experiment knobs
echo base; python tools/diag_hang.py 16384 8 6
echo team; ALTRO_B200_BACKWARD_TEAM=1 python tools/diag_hang.py 16384 8 6
echo qrc=0; ALTRO_B200_QRC_UNIFORM=0 python tools/diag_hang.py 16384 8 6
echo depth4; ALTRO_B200_FWD_DEPTH=4 python tools/diag_hang.py 16384 8 6
for ns in 2 4 16 32; do echo nsplit=$ns; python tools/diag_hang.py 16384 $ns 6; done
echo slots; python tools/diag_hang.py 16384 8 5; python tools/diag_hang.py 16384 8 4
