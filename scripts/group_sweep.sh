python scripts/group_tune.py "-;-" "-;-@4" "61;-@4" "-;-@3" "35;-@3" "-;-@2" "-;-@0,1,2" 2>&1 | grep "levels"
