"""TEST INFRASTRUCTURE ONLY.

``oracle`` holds the CPU restatement of the reference's arithmetic for SIDE's stereo hot
path.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product (``side_b200``)
never imports it and has no CPU fallback.
"""
