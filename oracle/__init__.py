"""CPU oracle for the PuzzleFusion++ denoise-and-verify hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``puzzlefusion-plusplus_b200/``) may import, call or link anything from this
directory.  The only legitimate users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker or the CPU baseline,
never as the thing measured or shipped.

What it is: a plain-PyTorch (fp32, CPU) functional restatement of the
reference's algorithm for SURVEY.md section 8(a) rows a1-a22.  First-party
arithmetic follows the cited reference lines; the un-vendored third-party
pieces (diffusers 0.21.4 Attention/FeedForward/DDPMScheduler, torch_cluster
fps, pytorch3d quaternion ops / normals, chamferdist) are restated from their
published algorithms in ``oracle/third_party.py``.

Parity pinning: the reference ships no tests, golden vectors or known-answer
fixtures for this path ("parity unpinned" by the reference itself).  The oracle
is therefore pinned against outputs of the reference's OWN first-party modules
executed in the build container: ``oracle/gen_golden.py`` imports
``/root/reference`` through the ``sys.modules`` shims in ``oracle/_shims.py``
(which provide only the missing third-party packages, implemented with
``oracle/third_party.py``), runs the reference modules on seeded inputs and
commits the results under ``tests/golden/``.  The third-party restatements
themselves remain unpinned (their sources are not in /root/reference and the
packages are not installable here); see DESIGN.md "Oracle".
"""
