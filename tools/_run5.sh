for f in 0 16; do echo "=== dbg flags $f"; PNB_PROF=1 PNB_DBG_FLAGS=$f timeout 200 python tools/tc_profile.py 2>&1 | grep "vs the default\|cycles per 128-row\|issuer: MMA\|loader\|wait weights"; done
