{-|
Module      : Qubism.StateVec
Description : drop-in replacement of src/Qubism/StateVec.hs over libqubism_sv.so

Same export list as the reference module (StateVec.hs:14-25).  The amplitudes live on the GPU;
a @StateVec n@ owns one @qb_state@ through a ForeignPtr whose finalizer is @qb_state_free@.
Pure functions (normalize, tensor, collapse) work on a lazy clone (qb_state_clone shares the device
shard until one of the two handles is written AND the other observed); so do the StateT ones
(measureQubit, measure), whose clone becomes the new state.  UNVERIFIED BY COMPILATION (no GHC in the build image).
-}
{-# LANGUAGE DataKinds, KindSignatures, ScopedTypeVariables, TypeOperators #-}
module Qubism.StateVec
  ( StateVec (UnsafeMkStateVec)
  , mkStateVec, mkStateVec', mkQubit
  , normalize, tensor, collapse, measureQubit, measure, dimension
  -- backend plumbing used by Qubism.QGate
  , withSV, cloneSV, context
  ) where

import GHC.TypeLits
import Data.Singletons
import Data.Singletons.TypeLits
import Data.Finite
import Control.Monad.Random
import Control.Monad.Trans.State.Strict
import Data.Complex
import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)
import Text.Printf

import Qubism.Algebra
import Qubism.CReg
import Qubism.Backend.FFI

newtype StateVec (n :: Nat) = UnsafeMkStateVec (ForeignPtr QbState)

-- | one context per process (device 0); created on first use
{-# NOINLINE context #-}
context :: Ptr QbCtx
context = unsafePerformIO $ alloca $ \pp -> check (c_qb_init 0 pp) >> peek pp

check :: IO CInt -> IO ()
check act = act >>= \rc -> if rc == 0 then pure () else error ("qubism_sv: status " ++ show rc)

wrap :: (Ptr (Ptr QbState) -> IO CInt) -> IO (StateVec n)
wrap f = alloca $ \pp -> check (f pp) >> peek pp >>= fmap UnsafeMkStateVec . newForeignPtr p_qb_state_free

withSV :: StateVec n -> (Ptr QbState -> IO a) -> IO a
withSV (UnsafeMkStateVec fp) = withForeignPtr fp

cloneSV :: StateVec n -> IO (StateVec n)
cloneSV sv = withSV sv $ \p -> wrap (c_qb_state_clone p)

qubits :: forall n a. (KnownNat n, Num a) => StateVec n -> a
qubits _ = fromIntegral $ fromSing (sing :: SNat n)

dimension :: forall (n :: Nat) a. (KnownNat n, Num a) => StateVec n -> a
dimension = qubits                                            -- StateVec.hs:74-75

mkStateVec :: forall n . KnownNat n => StateVec n             -- StateVec.hs:78-79
mkStateVec = mkStateVec' (sing :: Sing n)

mkStateVec' :: Sing n -> StateVec n                           -- StateVec.hs:83-85
mkStateVec' sn = unsafePerformIO . wrap $ c_qb_state_create context (fromIntegral (fromSing sn)) 1

mkQubit :: StateVec 1                                         -- StateVec.hs:88-89
mkQubit = mkStateVec

readAll :: KnownNat n => StateVec n -> IO [C]
readAll sv = let l = 2 ^ (qubits sv :: Int) in
  allocaArray l $ \buf -> withSV sv (\p -> check (c_qb_state_read p 0 (fromIntegral l) buf)) >> peekArray l buf

instance KnownNat n => Eq (StateVec n) where                  -- StateVec.hs:47-49
  a == b = unsafePerformIO $ do
    d <- cloneSV a
    withSV d $ \pd -> withSV b $ \pb -> check (c_qb_axpy_ri pd (-1) 0 pb)
    alloca $ \o -> withSV d (\pd -> check (c_qb_norm2 pd o)) >> ((< 0.000001) <$> peek o)

instance KnownNat n => VectorSpace (StateVec n) where         -- StateVec.hs:51-55
  zero = unsafePerformIO . wrap $ c_qb_state_create context (fromIntegral (natVal (Proxy :: Proxy n))) 0
  (zr :+ zi) .: a = unsafePerformIO $ do
    r <- cloneSV a; withSV r (\p -> check (c_qb_scale_ri p (realToFrac zr) (realToFrac zi))); pure r
  a +: b = unsafePerformIO $ do
    r <- cloneSV a; withSV r (\pr -> withSV b (\pb -> check (c_qb_axpy_ri pr 1 0 pb))); pure r
  neg a = unsafePerformIO $ do
    r <- cloneSV a; withSV r (check . c_qb_neg); pure r

instance KnownNat n => HilbertSpace (StateVec n) where        -- StateVec.hs:57-58
  a <.> b = unsafePerformIO . alloca $ \o ->
    withSV a (\pa -> withSV b (\pb -> check (c_qb_dotc pa pb o))) >> peek o

instance KnownNat n => Show (StateVec n) where                -- StateVec.hs:60-68
  show sv = concat . zipWith row [0 :: Integer ..] . unsafePerformIO $ readAll sv
    where row i z = printf "% 6.4f" (realPart z) ++ "  + " ++ printf "% 6.4f" (imagPart z) ++ "i"
                    ++ "  " ++ "|" ++ fmap (bit i) (take n [0 ..]) ++ ">\n"
          bit i j = if i `quot` 2 ^ (n - j - 1) `mod` 2 == 0 then '0' else '1'
          n = qubits sv :: Int

normalize :: StateVec n -> StateVec n                         -- StateVec.hs:91-92
normalize a = unsafePerformIO $ do
  r <- cloneSV a; withSV r (check . c_qb_normalize); pure r

tensor :: StateVec n -> StateVec m -> StateVec (n + m)        -- StateVec.hs:98-100
tensor a b = unsafePerformIO $ withSV a $ \pa -> withSV b $ \pb -> wrap (c_qb_tensor pa pb)

collapse :: forall n . KnownNat n => Finite n -> Bit -> StateVec n -> StateVec n   -- StateVec.hs:104-114
collapse i b sv = unsafePerformIO $ do
  r <- cloneSV sv
  withSV r $ \p -> check (c_qb_collapse p (fromIntegral (getFinite i)) (if b == One then 1 else 0))
  pure r

-- | StateVec.hs:118-129.  The draw stays in MonadRandom; the reduction, the rule
-- (One iff r < sqrt S1) and the collapse happen on the device.  The state read with 'get' stays
-- valid (value semantics): the measurement runs on a lazy clone (no copy unless the old value is
-- observed again), which becomes the new state.
measureQubit :: (MonadRandom m, KnownNat n) => Finite n -> StateT (StateVec n) m Bit
measureQubit i = do
  qr <- get
  r  <- getRandomR (0, 1 :: Double)
  let (bit, qr') = unsafePerformIO $ do
        w <- cloneSV qr
        b <- alloca $ \pb -> alloca $ \pp -> do
          withSV w $ \p -> check (c_qb_measure_qubit p (fromIntegral (getFinite i)) (realToFrac r) pb pp)
          peek pb
        pure (b, w)
  bit `seq` put qr'
  pure (if bit == 1 then One else Zero)

measure :: forall m n . (MonadRandom m, KnownNat n) => StateT (StateVec n) m CReg   -- StateVec.hs:133-137
measure = mkCReg <$> traverse measureQubit (take n [0 ..])
  where n = fromIntegral $ fromSing (sing :: Sing n)
