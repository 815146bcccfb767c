CMBPO_TC_DEBUG=2 timeout 120 python tools/poldbg.py 2>&1 | tail -22
CMBPO_TC_DEBUG=2 CMBPO_TC_TRACE_ONLY=1 timeout 120 python tools/poldbg.py 2>&1 | grep -A5 "trace stream 0" | head -8
