"""Pure-Python restatement (small cases only): field, hashes, challenger, Merkle, gates, proof codec, verifier."""
