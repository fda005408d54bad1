"""B200-native SRP-PHAT + shift-stack front end (drop-in for the hot path of
uw-x/AcousticSwarms-Speech's ``sep/Mic_Array.py``).  See DESIGN.md."""
__version__ = "0.1.0"
