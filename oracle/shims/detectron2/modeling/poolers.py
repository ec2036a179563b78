from oracle.upstream import ROIPooler, ROIAlign, convert_boxes_to_pooler_format, assign_boxes_to_levels  # noqa
