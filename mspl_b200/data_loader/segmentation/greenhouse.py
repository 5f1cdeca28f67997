"""Source-dataset -> greenhouse class tables (the constants of the reference's
data_loader/segmentation/greenhouse.py:14-58).  The kernels take them as <=256-entry uint8 tables."""
import numpy as np

GREENHOUSE_CLASS_LIST = ['end_of_plant', 'other_plant', 'artificial', 'ground', 'other']
NUM_GREENHOUSE_CLASSES = len(GREENHOUSE_CLASS_LIST)
IGNORE_LABEL = 4   # 'other'; also what merge_outputs writes for pixels without enough votes (uest_seg_multi_os.py:716)

# greenhouse ids: 1 other_plant, 2 artificial, 3 ground, 4 other.  No source class maps to 0 (end_of_plant).
_PLANT, _ARTIFICIAL, _GROUND, _OTHER = 1, 2, 3, 4

# CamVid (13): Sky Building Pole Road Pavement Tree SignSymbol Fence Car Pedestrian Bicyclist Road_marking Unlabeled
id_camvid_to_greenhouse = np.array([
    _OTHER, _ARTIFICIAL, _ARTIFICIAL, _GROUND, _GROUND, _PLANT, _ARTIFICIAL, _ARTIFICIAL, _ARTIFICIAL,
    _OTHER, _OTHER, _ARTIFICIAL, _OTHER])

# Cityscapes (19 + background): Road Sidewalk Building Wall Fence Pole TrafficLight TrafficSign Vegetation Terrain
# Sky Person Rider Car Truck Bus Train Motorcycle Bicycle Background
id_cityscapes_to_greenhouse = np.array(
    [_GROUND, _GROUND] + [_ARTIFICIAL] * 6 + [_PLANT, _GROUND] + [_OTHER] * 3 + [_ARTIFICIAL] * 6 + [_OTHER])

# Freiburg Forest (5): road grass tree sky obstacle
id_forest_to_greenhouse = np.array([_GROUND, _PLANT, _PLANT, _ARTIFICIAL, _ARTIFICIAL])

SOURCE_TABLES = {
    'camvid': id_camvid_to_greenhouse,
    'cityscapes': id_cityscapes_to_greenhouse,
    'forest': id_forest_to_greenhouse,
}
# class counts main() assigns to each source dataset name (uest_seg_multi_os.py:427-432)
SOURCE_NUM_CLASSES = {'camvid': 13, 'cityscapes': 20, 'forest': 5, 'greenhouse': 5}
