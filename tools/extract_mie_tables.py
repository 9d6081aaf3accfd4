#!/usr/bin/env python3
"""Extract the two 4096-entry Lorenz-Mie phase tables from the reference as DATA.

Reads the float literals of `mie[]` and `choppedMie[]` in
DeepestScatter_DataGen/src/Mie.cpp:8-8203 and writes them, as little-endian
float32, to deepestscatter_b200/data/mie_tables.f32 (mie first, then
choppedMie; 2*4096*4 = 32768 bytes).  Only the numeric data travels; the
normalisation / prefix-sum logic (Mie.cpp:8206-8282) is re-stated in
oracle/ and csrc/.  Run in the build container only (/root/reference is
not present on the GPU box).
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/DeepestScatter_DataGen/DeepestScatter_DataGen/src/Mie.cpp")
OUT = Path(__file__).resolve().parent.parent / "deepestscatter_b200" / "data" / "mie_tables.f32"


def main() -> int:
    text = REF.read_text(encoding="utf-8", errors="ignore")
    tables = {}
    for name in ("mie", "choppedMie"):
        m = re.search(r"float_t\s+%s\[\]\s*=\s*\{(.*?)\};" % name, text, re.S)
        if not m:
            raise SystemExit(f"table {name} not found")
        vals = re.findall(r"([-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)f", m.group(1))
        arr = np.array([float(v) for v in vals], dtype=np.float64).astype(np.float32)
        if arr.size != 4096:
            raise SystemExit(f"{name}: expected 4096 entries, got {arr.size}")
        tables[name] = arr
    mie, chopped = tables["mie"], tables["choppedMie"]
    # sanity values quoted in SURVEY.md section 8c
    assert abs(float(mie.astype(np.float64).mean()) - 5.2588) < 1e-3
    assert abs(float(chopped.astype(np.float64).mean()) - 0.52687) < 1e-4
    assert int(np.argmax(mie != chopped)) == 4081
    assert abs(float(mie[4095]) - 19086.0499712) < 1e-2
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.concatenate([mie, chopped]).astype("<f4").tofile(OUT)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
