# bash scripts/group_sweep.sh: the grouped cell launch under a few settings (edit the list; scripts/group_tune.py)
python scripts/group_tune.py "-;-" "-;-@4" "61;-@4" "-;-@3" "35;-@3" "-;-@2" "-;-@0,1,2" 2>&1 | grep "levels"
