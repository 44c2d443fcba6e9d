#!/bin/bash
# End-to-end smoke of the drop-in CLIs on a GPU box (tiny epoch counts, default precision, synthetic MREO-shape data).
set -e
cd "$(dirname "$0")/.."
( time python mr_gan.py --tables 1 --synthetic --seed 0 --epochs 2 ) 2>&1 | tail -6
( time python mr_gan.py --tables 5 6 --synthetic --seed 0 --epochs 1 ) 2>&1 | tail -6
( time python mr_gan.py --tables 3 --synthetic --seed 0 --epochs 1 ) 2>&1 | tail -6
( time python mr_nn.py --tables 2 --synthetic --seed 0 --epochs 2 ) 2>&1 | tail -6
