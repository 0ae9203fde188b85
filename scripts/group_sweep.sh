for sk in 0 1 2 3; do
echo "DBG_SKIP=$sk (1 = no epilogue loads, 2 = no epilogue stores)"
RSIS_B200_DBG_SKIP=$sk python scripts/group_tune.py "-;-@4" "61;-@4" "-;-@3" "35;-@3" "-;-" 2>&1 | grep "levels"
done
