"""`src.` namespace of the reference: its scripts and tests import `from src.channel_simulator import ...`
(test_phase1_transmission.py:8, test_phase2_ls.py:8-10, test_phase2_mmse.py:8-10) from the repository root.
With this package's directory on sys.path the same import lines resolve to the B200 drop-ins."""
