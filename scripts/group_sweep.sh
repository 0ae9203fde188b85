RSIS_B200_PRINT_PLAN=1 python scripts/group_tune.py "-;-" "-;-@0,1,2,3" "-;-@1,2,3,4" "-;-@2,3,4" "-;-@3,4" "-;-@0,1,2" "-;-@0,1" 2>&1 | grep "levels\|rsis group" | uniq
