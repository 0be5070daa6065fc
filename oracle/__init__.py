"""CPU oracle for the LLMRec hot path -- test infrastructure, never imported by llmrec_b200."""
