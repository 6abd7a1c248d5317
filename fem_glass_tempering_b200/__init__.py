"""B200-native SurroGlas hot path."""
