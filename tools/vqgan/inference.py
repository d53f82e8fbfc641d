"""tools/vqgan/inference.py — old path of fish_speech/models/dac/inference.py (wav <-> npy codec CLI)."""
from fish_speech_b200.models.dac.inference import load_model, main  # noqa: F401

if __name__ == "__main__":
    main()
