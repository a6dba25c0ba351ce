"""Re-export of the seeded synthetic-utterance generator (it lives on the product side,
telugu_asr_b200/synth.py, because bench.py needs it without importing the oracle)."""
from telugu_asr_b200.synth import make_waveforms, draw_lengths, DISTRIBUTIONS, to_pcm16  # noqa: F401
