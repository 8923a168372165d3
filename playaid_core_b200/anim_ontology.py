"""Static domain data the hot path needs (data only, no logic).

* `ACTIONS` -- the 63 action-class names in class-id order, i.e. `list(MOVE_TO_CLASS_ID.keys())`
  as built by reference playaid/anim_ontology.py:592-600 and handed to `CNNActionDetector`
  at playaid/ai_runner.py:164-167. The classifier's logits are indexed by this list.
* `STAGE_FOV` -- stage enum -> camera field of view in degrees, the `"fov"` column of
  `STAGE_ENUM_TO_DATA` (reference playaid/anim_ontology.py:497-570) that
  `Fighter.set_from_json` uses for the bbox projection (playaid/fighter.py:479-491).
  Unknown stage ids fall back to stage 0 there, and here.
* `FIGHTER_NAME_TO_ENUM` -- fighter display name -> the game's fighter-kind id, for the fighters the path's
  defaults, fixtures and synthetic logs name (a slice of reference playaid/anim_ontology.py:395-495;
  callers with other fighters pass their own mapping to `load_timeline_from_ai_output`).
"""

ACTIONS = [
    "Jab", "DashAttack", "ForwardTilt", "DownTilt", "UpTilt", "ForwardSmash", "DownSmash", "UpSmash",
    "NeutralSpecial", "ForwardSpecial", "DownSpecial", "UpSpecial", "NeutralAir", "ForwardAir", "BackAir",
    "DownAir", "UpAir", "ZAir", "Grab", "GrabRelease", "Parry", "Pummel", "ForwardThrow", "BackThrow",
    "DownThrow", "UpThrow", "Jump", "ShortHop", "Fall", "SpecialFall", "Shield", "ShieldStun", "ShieldDrop",
    "Damaged", "Wait", "Walk", "Squat", "Dash", "Run", "Turn", "PlatformDrop", "AirDodge", "Roll", "SpotDodge",
    "DownWait", "MissedTech", "TechInPlace", "TechRoll", "NormalGetUp", "GetUpAttack", "Taunt", "LedgeHang",
    "LedgeAttack", "LedgeNormalGetUp", "LedgeRoll", "LedgeJump", "LedgeGrab", "ItemPickup", "ItemThrow", "Slip",
    "Landing", "Undefined", "Grabbed",
]

MOVE_TO_CLASS_ID = {name: i for i, name in enumerate(ACTIONS)}

STAGE_FOV = {
    0: 50, 3: 50, 44: 50, 51: 50, 86: 50, 89: 50, 95: 30, 107: 50, 118: 50, 242: 50, 257: 50, 268: 50,
    293: 50, 295: 50, 330: 50, 347: 50, 351: 50, 361: 50,
}


def stage_fov(stage_id: int) -> int:
    return STAGE_FOV.get(stage_id, STAGE_FOV[0])

FIGHTER_NAME_TO_ENUM = {"Mario": 0, "Pikachu": 8, "Diddy Kong": 39, "Joker": 82, "Byleth": 86}
