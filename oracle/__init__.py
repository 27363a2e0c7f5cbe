"""CPU oracle for the Deep_Interpolation_Clustering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker (or as the CPU arm being timed), never as the thing
the GPU numbers are measured on.

Parity status: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned against *outputs of the
reference itself*: ``oracle/gen_golden.py`` imports the reference modules from
``/root/reference`` (with stub modules for its absent optional imports), runs
them on seeded synthetic inputs and commits inputs+outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every restatement
here against those fixtures.  The scikit-learn arithmetic behind the k-means
path is third-party and unpinned by the reference; fixtures record the
container's versions (scikit-learn 1.9.0, scipy 1.18.1).

Modules
-------
interp_oracle   numpy restatement of SingleChannelInterp / CrossChannelInterp /
                RBF forward + closed-form backward
dec_oracle      numpy restatement of ClusterAssignment / target_distribution /
                KL and their closed-form gradients
kmeans_oracle   numpy restatement of the Lloyd loop, the two "inertia"
                definitions and the gap statistic
ref_port        torch-CPU op-by-op port (autograd) used as the timed CPU arm
"""
