"""CPU oracle for the nunif hot paths (waifu2x tiled SR, iw3 depth->stereo).

TEST INFRASTRUCTURE ONLY.  Nothing under ``nunif_b200/`` may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only
as the checker / CPU baseline - never as the thing shipped or measured as the
product.

Every function is a plain torch-CPU / numpy fp32 restatement of the reference
algorithm and cites the reference file:line it follows (paths relative to
nagadomi/nunif @ d23721f).  Parity is *pinned*: ``oracle/gen_golden.py`` ran
the real reference (imported from /root/reference in the build container) on
seeded inputs and committed the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` holds the oracle to those vectors.
"""
