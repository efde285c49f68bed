"""Import-only stand-in for matplotlib (the reference imports pyplot in generate_primes.py but
never plots on the paths used here)."""
