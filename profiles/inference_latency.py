"""Inference latency at batch 1 / 64 (BASELINE.json configs[2]); prints bench.measure_inference()'s result."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import measure_inference  # noqa: E402

print(json.dumps(measure_inference(torch.device("cuda", 0), 256)))
