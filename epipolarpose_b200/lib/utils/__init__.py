"""Host-side helpers of the mirror: image / patch geometry, triangulation front ends, optimisers, augmentation."""
