#!/bin/bash
for r in 1 2 4 8; do echo "== R=$r"; AACLIP_HEAD_R=$r timeout 100 python tools/head_probe.py 64 2>&1 | grep "levels=4\|bfloat16 levels=1"; done
echo "== R=2 no PDL"; AACLIP_HEAD_R=2 AACLIP_HEAD_NOPDL=1 timeout 100 python tools/head_probe.py 64 2>&1 | grep "levels=4\|bfloat16 levels=1"
