"""pfpp-b200: B200-native denoise-and-verify engine for PuzzleFusion++ (hot path only)."""
__version__ = "0.1.0"
