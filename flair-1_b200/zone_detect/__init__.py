"""B200-native `flair-detect` (mirrors src/zone_detect/ of the reference for the hot path)."""
