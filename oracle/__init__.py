"""CPU oracle for the STFT-domain BSS update loop -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy (float64 / complex128) restatement of the update
algorithms of tky823/audio_source_separation (src/bss/{ilrma,iva,mnmf}.py,
src/algorithm/{nmf,projection_back}.py, src/utils/utils_linalg.py).  Every
function cites the reference file:line it follows.

It exists so that the CUDA path can be checked for parity.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product package
(`audio_source_separation_b200`) never imports it and has no CPU fallback.

Pinning status: the reference ships no golden vectors or numeric tests
(SURVEY.md section 4), so the oracle is pinned against the reference's own
classes executed in the build container (`oracle/pin/make_golden.py`, which
imports /root/reference/src behind a NumPy-1.x `linalg.solve` shim) and
against the known-answer loss values of SURVEY.md Appendix D.  The resulting
fixtures are committed under `tests/golden/`.
"""

from . import core, ilrma, auxiva, fastmnmf, nmf, synth  # noqa: F401  (mnmf, idlma: imported where used)
