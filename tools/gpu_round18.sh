#!/bin/bash
for sh in 0 1 2 3 4; do
echo "role_shift=$sh"
ALTRO_B200_ROLE_SHIFT=$sh python tools/diag_hang.py 16384 8 6
ALTRO_B200_ROLE_SHIFT=$sh python tools/diag_hang.py 16384 1 6
done
