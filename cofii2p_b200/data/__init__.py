"""Dataset front-ends adjacent to the hot path: the synthetic generator lives in
`cofii2p_b200.frames`; `kitti` reads the reference's on-disk KITTI npy layout."""
