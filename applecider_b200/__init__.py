"""applecider_b200 — B200-native (sm_100a) hot path of the AppleCiDEr multimodal classifier."""
__version__ = "0.1.0"
