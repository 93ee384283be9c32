"""CPU oracle for the DADD UNet-denoising hot path.

TEST INFRASTRUCTURE ONLY.  A self-contained fp32 PyTorch restatement of the
reference's algorithm (umutdundar99/progressive-stable-diffusion) for the path
named in BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``progressive_stable_diffusion_b200`` never does.

Pinning status
--------------
* ``oracle.processors``, ``oracle.purifier``, ``oracle.aoe`` are pinned against
  the *verbatim* reference modules (imported from /root/reference in the build
  container by ``tests/golden/make_golden.py``); the outputs are committed under
  ``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.
* ``oracle.unet`` and ``oracle.vae`` restate the graph of the un-vendored
  third-party dependency ``diffusers`` (>=0.31, ``pyproject.toml:27``; model id
  ``CompVis/stable-diffusion-v1-4``).  diffusers is not installed here and the
  reference holds no test touching that boundary: **parity unpinned** for those
  two modules.
* ``oracle.sampler`` restates ``src/pipelines/inference/inference_pipeline_ip.py``
  (not importable: needs omegaconf/lightning/diffusers); its constants are
  pinned by SURVEY.md Appendix B tables (tests/test_oracle_invariants.py).
"""
