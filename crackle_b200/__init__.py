"""crackle_b200: B200-native (sm_100a) implementation of seung-lab/crackle's per-z-slice compress / decompress
hot path behind the reference's own interface.  See DESIGN.md and INTEGRATION.md."""
import os as _os

# the z-chunk pipeline uses up to 2 streams per chunk: give them their own hardware work queues (read at CUDA context creation)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .codec import (Context, compress, decompress, decompress_binary_image, decompress_range, default_context,  # noqa: F401
                    header, z_range_for_label, voxel_counts, centroids, bounding_boxes, labels, num_labels, contains, reencode, zstack, zsplit, zshatter, voxel_connectivity_graph)

__all__ = ["Context", "compress", "decompress", "decompress_binary_image", "decompress_range", "default_context", "header",
           "z_range_for_label", "voxel_counts", "centroids", "bounding_boxes", "labels", "num_labels", "contains", "reencode", "zstack", "zsplit", "zshatter", "voxel_connectivity_graph"]
