"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the B200 monotonic-alignment path.  Importable only from
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  ``aligner_b200`` never imports this package; the
product path has no CPU fallback.

Contents
  mas_oracle.c   plain-C restatement of core.pyx:7-45 (see its header)
  mas.py         ctypes/numpy wrappers + loader for the compiled reference in _ref/
  neg_cent.py    fp64 restatement of the Glow-TTS / VITS / OTA score matrices
                 (parity UNPINNED: the reference snapshot holds no such code)
  build.py       compiles mas_oracle.c and, when /root/reference exists, the
                 reference's own core.pyx into oracle/_ref/
  make_golden.py regenerates tests/golden/ from the unmodified reference
"""
