"""CPU oracle (test infrastructure only) — see cvoracle.h."""
