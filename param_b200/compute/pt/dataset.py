"""Embedding shapes of the reference's `driver.py emb --dataset A|B`
(train/compute/pt/dataset.py:58-85): tuples of (features, embdim, nnz, batch)."""

_A_BATCHES = [512 << k for k in range(8)]            # 512 .. 65536
emb_A = [(rows, 128, 30, b) for rows in (14_000_000, 26_000_000) for b in _A_BATCHES]
emb_B = [(4_800_000, 56, 34, 2048 << k) for k in range(6)]  # 2048 .. 65536

# config 1 of BASELINE.json (plumbing case) and the per-table shape of config 2
emb_cfg1 = [(1_000_000, 64, 20, 512)]
emb_cfg2_table = [(10_000_000, 128, 20, 65536)]
