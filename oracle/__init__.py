"""CPU oracle — TEST INFRASTRUCTURE ONLY (parity unpinned). See oracle/oracle.py and oracle/prosody_oracle.c.

Nothing under prosody-control-french-tts_b200/ may import this package."""
