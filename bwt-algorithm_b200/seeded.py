"""FM / k-mer seeded seed-and-extend (bwt.py:2027-2095, 2562-2695) -- see module body."""
