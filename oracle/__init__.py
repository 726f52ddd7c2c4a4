"""CPU oracle of the xhistogram hot path — test infrastructure, never imported by the product."""
