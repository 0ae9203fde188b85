# bash scripts/group_sweep.sh: the grouped cell launch of every wavefront composition of the skewed schedule (scripts/group_tune.py)
RSIS_B200_PRINT_PLAN=1 python scripts/group_tune.py "-;-@0" "-;-@0,1" "-;-@0,1,2" "-;-@0,1,2,3" "-;-" "-;-@1,2,3,4" "-;-@2,3,4" "-;-@3,4" "-;-@4" 2>&1 | grep "levels\|rsis group" | uniq
