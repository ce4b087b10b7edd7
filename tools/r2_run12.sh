#!/bin/bash
python tools/split_sweep.py cornell_monkey 32 2>&1 | tail -4
python tools/split_sweep.py cornell_boxes 32 2>&1 | tail -4
python tools/split_sweep.py matball 8 2>&1 | tail -4
