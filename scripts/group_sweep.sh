RSIS_B200_PRINT_PLAN=1 python scripts/group_tune.py "-;-" "8,22,22,32,64;-" "8,22,22,30,66;-" "8,22,22,34,62;-" "8,22,25,31,62;-" "-;-@1,2,3,4" "-;-@0,1,2,3" 2>&1 | grep "levels\|rsis group" | uniq
