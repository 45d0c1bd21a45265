#!/bin/bash
# bicycle-only experiment build: schedule timings for a few speculation widths
for sl in 3 4 6; do
python tools/diag_hang.py 16384 8 $sl
done
python tools/diag_hang.py 16384 1 3
ALTRO_B200_INLINE_DERIV=0 python tools/diag_hang.py 16384 8 4
