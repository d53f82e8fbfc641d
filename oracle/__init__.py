"""ORACLE — test infrastructure only.  CPU restatement of the reference hot path + the recipe that
pins it against the real reference (see oracle/lm_oracle.py, oracle/codec_oracle.py, oracle/make_golden.py).
The product package (fish_speech_b200) never imports this."""
