"""CPU oracle for the OmniFusion tangent-patch inference path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch-CPU / numpy, fp32)
restatement of the reference algorithm, used as the checker for the CUDA
product in ``omnifusion_b200``.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  Nothing under ``omnifusion_b200/`` imports it, and the product path fails
loudly when its CUDA library is missing - there is no CPU fallback.

Parity pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself*, produced in the authoring container by importing /root/reference
(``tests/golden/make_golden.py``) and committed as small fixtures under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle function
against those fixtures.

Third-party arithmetic the reference relies on (not vendored under
/root/reference; versions are unpinned in its requirements.txt): PyTorch ATen
CPU kernels (grid_sample, conv, batch_norm, interpolate, softmax, layer_norm,
gelu, normalize), torchvision's ResNet-34 topology and numpy sin/cos.  The
oracle calls the same torch-CPU / numpy primitives (torch 2.11.0 here), so the
restated op sequence - not a re-derivation of those primitives - is what is
being pinned.
"""
