#!/bin/bash
# r2ah: resident optimize() at 1920x1080, 6 neighbours: grey bytes vs colour float views
O=gpurun_out/r2ah; mkdir -p $O
python benchmarks/optimize_resident.py --reps 3 > $O/grey.json 2> $O/grey.err
python benchmarks/optimize_resident.py --reps 3 --colour > $O/colour.json 2> $O/colour.err
python benchmarks/optimize_resident.py --reps 3 --shading > $O/grey_S.json 2> $O/grey_S.err
cut -c1-700 $O/grey.json; cut -c1-700 $O/colour.json; cut -c1-700 $O/grey_S.json; tail -2 $O/colour.err
