"""ModelManager / FocoosModel facade (plumbing, no GPU): registry lookup, config validation, family dispatch."""
import pytest

from focoos_b200 import BisenetFormer, FAIDetr, FAIMaskFormer, ModelManager
from focoos_b200.processor import DETRProcessor, MaskFormerProcessor


@pytest.mark.parametrize("name,cls,proc", [("fai-detr-l-obj365", FAIDetr, DETRProcessor), ("fai-mf-l-coco-ins", FAIMaskFormer, MaskFormerProcessor),
                                            ("bisenetformer-l-ade", BisenetFormer, MaskFormerProcessor)])
def test_get_builds_the_right_family(name, cls, proc):
    fm = ModelManager.get(name)
    assert isinstance(fm.model, cls) and type(fm.processor) is proc and not fm.model.training


def test_config_override_and_validation():
    fm = ModelManager.get("fai-detr-l-coco", num_classes=3)
    assert fm.model.num_classes == 3 and fm.model.state_dict()["head.predictor.enc_score_classifier.weight"].shape[0] == 3
    with pytest.raises(ValueError):
        ModelManager.get("fai-detr-l-coco", not_a_field=1)
    with pytest.raises(ValueError):
        ModelManager.get("no-such-model")
