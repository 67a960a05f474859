"""CPU oracle for the UniMedVL unified forward path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (plain torch CPU
tensor ops, fp32 math with the reference's bf16 rounding points made explicit)
of the algorithm implemented by the reference under
``codes/modeling/unimedvl/{bagel,qwen2_navit,siglip_navit,modeling_utils}.py``,
``codes/modeling/qwen2/modeling_qwen2.py`` and ``codes/modeling/autoencoder.py``.
Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline -- never from the product package
``unimedvl_b200`` (which fails loudly when its CUDA library is missing).

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``
(committed, with the fixtures it wrote under ``tests/golden/``).  The fixtures
were produced on CPU, where ``torch.autocast("cpu")`` leaves LayerNorm/GroupNorm/
``torch.norm`` in bf16; the reference on CUDA computes those in fp32
(SURVEY.md section 8a).  ``Semantics.cpu`` reproduces the CPU fixtures bit-for-bit
in structure; ``Semantics.cuda`` (what GPU users of the reference see, and what
the engine implements) differs only at those documented points.
"""

from .numerics import Semantics  # noqa: F401
