{-|
Module      : Qubism.QGate
Description : drop-in replacement of src/Qubism/QGate.hs over libqubism_sv.so

Same export list and the same type signatures as the reference module (QGate.hs:14-31).  A
@QGate n@ is SYMBOLIC: its width plus a sum of coefficient * product-of-factors, each factor a
(multi-)controlled 1-qubit matrix or a small dense block; @(#>)@ streams the factors through the C
ABI where they are fused into few passes over HBM.  Mirrors qubism_b200/qgate.py, which is the
tested implementation of the same algebra (tests/test_gpu_parity.py drives it against the oracle).
UNVERIFIED BY COMPILATION (no GHC in the build image).

@(#>)@ is pure: it takes a LAZY clone (qb_state_clone: a second handle on the same device shard, no
copy, no flush) and queues on that, so the interpreter's one-@#>@-per-primitive-op pattern
(QASM/Simulation.hs:94-122) still ends up as a few fused passes.
-}
{-# LANGUAGE DataKinds, KindSignatures, ScopedTypeVariables, TypeOperators #-}
module Qubism.QGate
  ( QGate, (#>), gate, ident, pauliX, pauliY, pauliZ, hadamard, unitary
  , cnot, controlled, ifBit, kronecker, onJust, onEvery, onRange
  ) where

import GHC.TypeLits
import Data.Singletons
import Data.Singletons.TypeLits
import Data.Finite
import Data.Bits (shiftL, shiftR, testBit, (.&.), (.|.), complement)
import Data.Complex
import Data.List (nub, sort, foldl')
import Control.Monad.Trans.State.Strict
import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)
import           Numeric.LinearAlgebra ((><))
import qualified Numeric.LinearAlgebra as LA

import Qubism.Algebra
import Qubism.StateVec
import Qubism.CReg
import Qubism.Backend.FFI

-- | targets (first = most significant index bit of the matrix), row-major matrix, controls
data Factor = Factor [Int] [C] [Int]
-- | a term: coefficient and factors in APPLICATION order
type Term = (C, [Factor])
-- | width (the value-level copy of n: 'kronecker' and 'Eq' carry no KnownNat in the reference's
-- signatures) and the sum of terms
data QGate (n :: Nat) = UnsafeMkQGate Int [Term]

shiftF :: Int -> Factor -> Factor
shiftF d (Factor qs m cs) = Factor (map (+ d) qs) m (map (+ d) cs)

-- ---- the small dense form (host side; Eq, and controlled on a gate that touches its control) ----
-- | dense 2^w x 2^w matrix of one factor, row-major, qubit 0 = most significant index bit
factorDense :: Int -> Factor -> [[C]]
factorDense w (Factor qs m cs) = [ [ entry r c | c <- [0 .. dim - 1] ] | r <- [0 .. dim - 1] ]
  where
    dim = 1 `shiftL` w :: Int
    k   = length qs
    bitOf x q = (x `shiftR` (w - 1 - q)) .&. 1
    sub x = foldl' (\a q -> (a `shiftL` 1) .|. bitOf x q) 0 qs
    rest x = foldl' (\a q -> a .&. complement (1 `shiftL` (w - 1 - q))) x qs
    active c = all (\q -> bitOf c q == 1) cs
    entry r c
      | not (active c)     = if r == c then 1 else 0
      | rest r /= rest c   = 0
      | otherwise          = m !! (sub r * (1 `shiftL` k) + sub c)

mmul :: [[C]] -> [[C]] -> [[C]]
mmul a b = [ [ sum (zipWith (*) row col) | col <- cols ] | row <- a ]
  where cols = foldr (zipWith (:)) (repeat []) b

identD :: Int -> [[C]]
identD w = [ [ if r == c then 1 else 0 | c <- [0 .. d - 1] ] | r <- [0 .. d - 1 :: Int] ] where d = 1 `shiftL` w

-- | sum_k coef_k * (f_last ... f_first)
dense :: Int -> [Term] -> [[C]]
dense w ts = foldl' (zipWith (zipWith (+))) zeroD [ map (map (c *)) (foldl' (\acc f -> factorDense w f `mmul` acc) (identD w) fs) | (c, fs) <- ts ]
  where zeroD = map (map (const 0)) (identD w)

toLA :: Int -> [[C]] -> LA.Matrix C
toLA w rows = (d><d) (concat rows) where d = 1 `shiftL` w

-- | "Close enough" equality testing (QGate.hs:54-56): spectral norm of the difference, small n only
instance Eq (QGate n) where
  UnsafeMkQGate w a == UnsafeMkQGate _ b = LA.norm_2 (toLA w (dense w a) - toLA w (dense w b)) < 0.000001

instance KnownNat n => Semigroup (QGate n) where              -- QGate.hs:58-59: (a <> b) #> v = a #> (b #> v)
  UnsafeMkQGate w a <> UnsafeMkQGate _ b = UnsafeMkQGate w [ (ca * cb, fb ++ fa) | (ca, fa) <- a, (cb, fb) <- b ]
instance KnownNat n => Monoid (QGate n) where mempty = ident  -- QGate.hs:61-62
instance KnownNat n => VectorSpace (QGate n) where            -- QGate.hs:64-68
  zero = UnsafeMkQGate (fromIntegral (natVal (Proxy :: Proxy n))) []
  z .: UnsafeMkQGate w a = UnsafeMkQGate w [ (z * c, f) | (c, f) <- a ]
  UnsafeMkQGate w a +: UnsafeMkQGate _ b = UnsafeMkQGate w (a ++ b)
  neg (UnsafeMkQGate w a) = UnsafeMkQGate w [ (negate c, f) | (c, f) <- a ]
instance KnownNat n => Algebra (QGate n) where (*:) = (<>)    -- QGate.hs:70-71

g1 :: [C] -> QGate 1
g1 m = UnsafeMkQGate 1 [(1, [Factor [0] m []])]

ident :: forall n . KnownNat n => QGate n                     -- QGate.hs:86-87
ident = UnsafeMkQGate (fromIntegral (natVal (Proxy :: Proxy n))) [(1, [])]
pauliX, pauliY, pauliZ, hadamard :: QGate 1                   -- QGate.hs:90-108
pauliX   = g1 [0, 1, 1, 0]
pauliY   = g1 [0, 0 :+ (-1), 0 :+ 1, 0]
pauliZ   = g1 [1, 0, 0, -1]
hadamard = g1 (map (/ sqrt 2) [1, 1, 1, -1])

unitary :: Double -> Double -> Double -> QGate 1              -- QGate.hs:112-118 (the formula IS the specification)
unitary theta phi lambda = g1 [a, b, c, d]
  where a =  cis (phi+lambda/2) * ( cos (theta/2) :+ 0 )
        b = -cis (phi-lambda/2) * ( sin (theta/2) :+ 0 )
        c =  cis (phi-lambda/2) * ( sin (theta/2) :+ 0 )
        d =  cis (phi+lambda/2) * ( cos (theta/2) :+ 0 )

onJust :: forall n . KnownNat n => Finite n -> QGate 1 -> QGate n          -- QGate.hs:148-154
onJust i (UnsafeMkQGate _ ts) =
  UnsafeMkQGate (fromIntegral (natVal (Proxy :: Proxy n))) [ (c, map (shiftF (fromIntegral (getFinite i))) fs) | (c, fs) <- ts ]

onRange :: forall n . KnownNat n => Finite n -> Finite n -> QGate 1 -> QGate n   -- QGate.hs:164-165
onRange f l m = mconcat $ map (\i -> onJust i m) [f..l]

onEvery :: forall n . KnownNat n => QGate 1 -> QGate n                     -- QGate.hs:158-160
onEvery m = mconcat [ onJust (finite i) m | i <- [0 .. natVal (Proxy :: Proxy n) - 1] ]

-- | QGate.hs:142-144: a on the first qubits, b on the rest (the widths travel with the values)
kronecker :: QGate n -> QGate m -> QGate (m+n)
kronecker (UnsafeMkQGate wa a) (UnsafeMkQGate wb b) =
  UnsafeMkQGate (wa + wb) [ (ca * cb, map (shiftF wa) fb ++ fa) | (ca, fa) <- a, (cb, fb) <- b ]

-- | QGate.hs:125-132: M.P + I - P with P = diag(bit_i).  For a single product whose factors do not
-- touch qubit i this is the ordinary controlled gate (controls compose, nesting gives
-- multi-controlled gates).  Otherwise -- a sum, or a gate acting on its own control -- the literal
-- formula is evaluated on the few qubits involved and applied as ONE dense block (qb_apply_kq).
controlled :: forall n . KnownNat n => Finite n -> QGate n -> QGate n
controlled i (UnsafeMkQGate w ts)
  | [(1, fs)] <- ts, all free fs = UnsafeMkQGate w [(1, map addC fs)]
  | otherwise                    = UnsafeMkQGate w [(1, [Factor qs (concat block) []])]
  where
    k = fromIntegral (getFinite i) :: Int
    free (Factor fq _ fc) = k `notElem` fq && k `notElem` fc
    addC (Factor fq m fc) = Factor fq m (fc ++ [k])
    -- the qubits the gate involves, plus the control, renumbered 0..kk-1 in ascending order
    qs  = sort . nub $ k : concat [ fq ++ fc | (_, fs) <- ts, Factor fq _ fc <- fs ]
    kk  = length qs
    pos q = length (takeWhile (/= q) qs)
    small = dense kk [ (c, [ Factor (map pos fq) m (map pos fc) | Factor fq m fc <- fs ]) | (c, fs) <- ts ]
    pbit j = if testBit (j :: Int) (kk - 1 - pos k) then 1 else 0 :: C
    -- (M . P + I - P)[r][c] = M[r][c] * p_c + delta_rc * (1 - p_c), entry by entry as the reference
    -- forms it: (m <> projection) + ident - projection
    block = [ [ (small !! r !! c) * pbit c + (if r == c then 1 else 0) - (if r == c then pbit c else 0)
              | c <- [0 .. (1 `shiftL` kk) - 1] ] | r <- [0 .. (1 `shiftL` kk) - 1] ]

cnot :: KnownNat n => Finite n -> Finite n -> QGate n                      -- QGate.hs:121-122
cnot c t = controlled c . onJust t $ pauliX

ifBit :: KnownNat n => Bit -> QGate n -> QGate n                           -- QGate.hs:136-137
ifBit b g = if (b == One) then g else ident

emit :: Ptr QbState -> Factor -> IO ()
emit p (Factor [q] m []) = withArray m $ \pm -> ck (c_qb_apply_1q p (fromIntegral q) pm)
emit p (Factor [q] m cs) = withArray m $ \pm -> withArrayLen (map fromIntegral cs) $ \nc pc ->
  ck (c_qb_apply_ctrl_1q p pc (fromIntegral nc) (fromIntegral q) pm)
emit p (Factor qs m cs)  = withArray m $ \pm -> withArrayLen (map fromIntegral qs) $ \k pq ->
  withArrayLen (map fromIntegral cs) $ \nc pc -> ck (c_qb_apply_kq p pq (fromIntegral k) pm pc (fromIntegral nc))

ck :: IO CInt -> IO ()
ck act = act >>= \rc -> if rc == 0 then pure () else error ("qubism_sv: status " ++ show rc)

-- | one product term on a lazy clone of sv
applyTerm :: Term -> StateVec n -> IO (StateVec n)
applyTerm (c, fs) sv = do
  r <- cloneSV sv
  withSV r $ \p -> do
    mapM_ (emit p) fs
    if c == 1 then pure () else ck (c_qb_scale_ri p (realToFrac (realPart c)) (realToFrac (imagPart c)))
  pure r

infixr 5 #>
(#>) :: QGate n -> StateVec n -> StateVec n                                -- QGate.hs:78-80 (pure)
(#>) (UnsafeMkQGate _ [t]) sv = unsafePerformIO (applyTerm t sv)
(#>) (UnsafeMkQGate _ ts) sv = unsafePerformIO $ do                        -- A v + B v + ... (zero gate: 0 * v)
  acc <- applyTerm (0, []) sv
  mapM_ (\t -> applyTerm t sv >>= \tv -> withSV acc (\pa -> withSV tv (\pt -> ck (c_qb_axpy_ri pa 1 0 pt)))) ts
  pure acc

gate :: Monad m => QGate n -> StateT (StateVec n) m ()                      -- QGate.hs:83-84
gate g = state $ \qr -> ((), g #> qr)
