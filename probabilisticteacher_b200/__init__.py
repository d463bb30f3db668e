"""probabilisticteacher_b200: B200-native hot path of Probabilistic Teacher."""
__version__ = "0.1.0"
