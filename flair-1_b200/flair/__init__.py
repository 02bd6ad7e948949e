"""B200-native `flair` predict + metrics path (mirrors src/flair/ of the reference for that path)."""
