#!/bin/bash
python tools/diag_overlap.py 2>&1 | grep -v Warn | tail -14
echo ---- after a graph capture in the same process
python tools/diag_overlap.py graph 2>&1 | grep -v Warn | tail -14
