for i in 1 2 3 4; do echo "=== run $i"; timeout 200 python tools/debug_train2.py 2>&1 | grep -E "rel|Error|error" ; done
