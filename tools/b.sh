#!/bin/bash
# build libhdpo_b200.so from anywhere, show only the persistent kernel's ptxas lines
cd /root/repo && python -m neural_inventory_control_b200.build 2>&1 | grep -v "small\|sym\|warehouse_head" | tail -${1:-4}
