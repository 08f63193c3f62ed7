"""argus_b200: B200-native (sm_100a) implementation of the pculbertson/argus training/inference hot path."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
