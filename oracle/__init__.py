"""CPU oracle (test infrastructure only).  See oracle/whisper_eeg.py for the rules on who may import this."""
