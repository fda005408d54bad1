"""CPU oracle for the SRP-PHAT + shift-stack hot path of uw-x/AcousticSwarms-Speech.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and there only
as the checker (or the CPU baseline being timed), never as the thing shipped.
The product path (``acousticswarms_speech_b200``) never imports ``oracle`` and
fails loudly when its CUDA library is missing.

Contents
--------
``pra_stft``      restatement of ``pyroomacoustics.transform.stft.analysis``
                  (third-party, pinned ``pyroomacoustics==0.5.0`` in the
                  reference's requirements.txt:10; NOT vendored, NOT installable
                  here) -- assumption A1, **parity unpinned** for this one
                  function (no reference test or fixture pins it).
``srp_oracle``    numpy restatement of SRP_PHAT scoring
                  (sep/Traditional_SP/SRP_Prunning.py:221-243, 368-434).
``prune_oracle``  numpy restatement of the pruning
                  (SRP_Prunning.py:19-61, 347-357, 500-643) and of Patch
                  (sep/Traditional_SP/Patch_3D.py:3-93).
``geometry_oracle`` literal restatement of the hypercube table build
                  (SRP_Prunning.py:102-246, 257-344).
``shift_oracle``  roll_by_gather / shift loop / normalize_input
                  (sep/training/JointModel/network.py:12-25, 75-83;
                  sep/training/SpeakerLocalization/network.py:28-47).
``subdivide_oracle`` search_area / binary_area_divide_width /
                  binary_search_baseline (sep/helpers/local_utils_3d.py:13-17,
                  212-388).
``spotform_oracle`` everything Spotform_Small_Patch_Parallel does after the network
                  (sep/Mic_Array.py:32-81, 267-383) and Clustering_new (:18-28, 399-500,
                  sep/helpers/eval_utils.py:43-82).
``fake_librosa``  deterministic stand-in for the two librosa calls of split_wav
                  (librosa is absent here); used for BOTH the reference run that
                  makes the fixtures and the tests -- it pins everything of
                  Clustering_new except librosa itself (**unpinned**, like A1).
``ref_loader``    imports the UNMODIFIED reference from /root/reference under
                  inert stubs for its absent third-party packages (only works in
                  the build container; used to pin the restatements and to
                  generate tests/golden/*.npz via ``make_golden.py``).

Pinning status: the reference ships no golden vectors, KATs or tests
(SURVEY.md section 4).  Every restatement here is pinned against outputs of the
reference's own classes run in the build container (tests/golden/, generated
by oracle/make_golden.py, and re-checked live by tests/test_oracle_vs_reference.py
whenever /root/reference is present).  The single unpinned link is A1 above.
"""
