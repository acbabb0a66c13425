# see optik_b200/csrc/ldl6.cuh (generated); kept for provenance -- run from the repo root to regenerate
