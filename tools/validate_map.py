#!/usr/bin/env python
"""Validate Pinball .cfg maps (needs a CUDA device: the map object lives on the GPU).
    python tools/validate_map.py maps/easy.cfg maps/hard.cfg"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skill_chaining_with_graphs_b200 as scg  # noqa: E402

bad = 0
for path in sys.argv[1:]:
    m = scg.PinballMap.from_file(path)
    probs = m.validate()
    print(f"{path}: {m.n_edges} edges, {len(m.polygons)} polygons, {m.n_candidates} grid candidates: "
          + ("ok" if not probs else "; ".join(probs)))
    bad += bool(probs)
sys.exit(1 if bad else 0)
