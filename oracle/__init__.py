"""CPU oracle for the walrus_b200 hot path — TEST INFRASTRUCTURE, never imported by the product."""
